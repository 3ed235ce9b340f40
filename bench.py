"""Benchmark of the acoustic front-end hot path (BASELINE.json metric: fbank audio-sec/sec, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

A "step" is one pass of the hot path over one batch of synthetic 16 kHz int16 audio:
BASELINE config[1] -- AISHELL front-end + online speed perturb {0.9, 1.0, 1.1} + per-utterance
normalisation (the recipe default) + SpecAugment (3, 2, 50, 10) + global CMVN, batch 256 of 2-10 s
utterances -- plus accumulation of the CMVN statistics.  With N > 1 every rank runs the same
workload on its own utterances (weak scaling, no data-path collective) and the one collective of the
path, the 161-double CMVN-stats all-reduce, closes the timed region.

Keys of the JSON line (see the task contract):
  value        device-resident whole-job throughput: PCM already in HBM, batches prepared (oe_batch_prepare), the timed
               region holds the four kernel launches per step (metadata fetch, descriptors, fbank, in-place completion);
  e2e          the same metric through the public collate API from pinned HOST memory, H2D inside the timed region;
               `value` keeps the features on the GPU (collate for a GPU trainer), `to_host` also brings the padded
               feature tensor back to pinned host memory -- the reference's own boundary (CPU tensors, dataset.py:232-238);
               `h2d_ceiling_gbs`: plain cudaMemcpyAsync from pinned memory on every rank at once, measured in this run;
  roofline     the dominant kernel as it runs INSIDE the timed step against the measured HBM peak;
  configs      the other BASELINE configurations (device-resident): config 1 shape, 3 (LibriSpeech shape, one list cut
               by shard_by_length, CMVN statistics + one all-reduce per pass), 4 (short utterances + spec_sub),
               5 (16-frame streaming windows);
  cpu_baseline the reference CPU path (oracle collate port calling torchaudio.compliance.kaldi.fbank) on the host cores.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ('configs[1]: AISHELL front-end, batch 256 x U[2,10] s 16 kHz int16, online speed perturb '
            '{0.9,1.0,1.1}, per-utt norm, spec_aug(3,2,50,10), global CMVN + CMVN-stats accumulation')
BATCH = 256
POOL = 8            # distinct batches cycled so the inputs (~0.4 GB) exceed the 126 MB L2
CONF = {'resample_rate': 16000, 'speed_perturb_rate': 0, 'speeds': [0.9, 1.1, 0.1], 'wav_dither': 0.0,
        'mel_bins': 80}
AUG = dict(num_t_mask=3, num_f_mask=2, max_t=50, max_f=10)
SUB = dict(num_t_sub=3, max_t=30)
SPEEDS = (0.9, 1.0, 1.1)
JSON_OUT = sys.stdout


def workload(rank):
    """Lengths (seed 1002), speeds (seed 1003) -- BASELINE.md section 4; per-rank offsets for weak scaling."""
    rng_l = np.random.default_rng(1002 + 7919 * rank)
    rng_s = np.random.default_rng(1003 + 7919 * rank)
    lens = np.round(rng_l.uniform(2.0, 10.0, BATCH) * 16000).astype(np.int32)
    speeds = [SPEEDS[i] for i in rng_s.integers(0, 3, BATCH)]
    return lens, speeds


def synth_pool_host(lens, rank, count):
    """int16 Gaussian sigma=3000 (clipped) packed batches in pinned host memory."""
    import torch
    from openeat_b200.frontend import aligned_offsets
    offs, total = aligned_offsets(lens)
    gen = torch.Generator().manual_seed(1001 + rank)
    pool = []
    for _ in range(count):
        x = (torch.randn(total, generator=gen) * 3000.0).round_().clamp_(-32768, 32767).to(torch.int16)
        pool.append(x.pin_memory())
    return pool, offs


def synth_dev(total, dev, seed):
    """int16 Gaussian sigma=3000 (clipped), generated on the device."""
    import torch
    gen = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(total, dtype=torch.int16, device=dev)
    step = 1 << 26
    for a in range(0, total, step):
        n = min(step, total - a)
        out[a:a + n] = (torch.randn(n, generator=gen, device=dev) * 3000.0).round_().clamp_(-32768, 32767).to(torch.int16)
    return out


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference CPU path (also the cpu_baseline leg)
_CPU = {}


def _cpu_init(pcm, lens, offs, speeds):
    import torch
    torch.set_num_threads(1)
    _CPU.update(pcm=pcm, lens=lens, offs=offs, speeds=speeds)
    from oracle import collate
    _CPU['collate'] = collate.AudioCollate(feature_extraction_conf=CONF, normalization=True, spec_aug=True,
                                           spec_aug_conf=AUG, speed_fn=collate.default_speed_fn())


def _cpu_work(idx):
    """The reference's per-utterance chain on one core: speed perturb -> kaldi.fbank -> _normalization ->
    _spec_augmentation (oracle/collate.py restates dataset.py:39-118,185-209)."""
    pcm, lens, offs, speeds = _CPU['pcm'], _CPU['lens'], _CPU['offs'], _CPU['speeds']
    batch = [('u%d' % i, (pcm[offs[i]:offs[i] + lens[i]].astype(np.float32), 16000), [1], speeds[i]) for i in idx]
    random.seed(1234 + int(idx[0]))
    _, out = _CPU['collate'](batch)
    return float(sum(lens[i] for i in idx)) / 16000.0, int(out['features_length'].sum())


def cpu_reference_throughput(lens, speeds, steps, warmup):
    """Audio-sec/sec of the reference CPU pipeline on all host cores; each step = the same 256-utterance
    batch split over `cores` worker processes (16-utterance sub-batches, like the recipe's batch 16)."""
    import multiprocessing as mp
    from openeat_b200.frontend import aligned_offsets
    cores = len(os.sched_getaffinity(0))
    offs, total = aligned_offsets(lens)
    rng = np.random.default_rng(1001)
    pcm = np.clip(np.round(rng.normal(0.0, 3000.0, total)), -32768, 32767).astype(np.int16)
    from oracle import collate
    impl = collate.default_fbank_fn()[1]
    chunks = [list(range(i, min(i + 16, len(lens)))) for i in range(0, len(lens), 16)]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(pcm, lens, offs, speeds)) as pool:
        for _ in range(warmup):
            pool.map(_cpu_work, chunks)
        t0 = time.perf_counter()
        secs = 0.0
        for _ in range(steps):
            secs += sum(r[0] for r in pool.map(_cpu_work, chunks))
        dt = time.perf_counter() - t0
    return secs / dt, cores, dt / steps * 1e3, impl


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on the host cores: same config, metric and unit as our arm, the
    requested --steps / --warmup (one step = one 256-utterance batch, ~0.1 s on 16 cores; capped at 50 steps)."""
    if rank != 0:
        return
    lens, speeds = workload(0)
    steps, warmup = max(1, min(args.steps, 50)), max(1, min(args.warmup, 10))
    value, cores, ms, impl = cpu_reference_throughput(lens, speeds, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'fbank_audio_seconds_per_second', 'value': value, 'unit': 'audio-s/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'note': 'CPU reference path, rank 0 only'},
        'cpu_baseline': {'value': value, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
                         'sample': 'one 256-utterance batch (%.0f audio-s) per step, %d steps; oracle collate port '
                                   'calling %s (the function the reference calls at dataset.py:93-100), torchaudio.functional.speed '
                                   '(libsox substitute), numpy norm/spec_aug; %d worker processes x 1 thread'
                                   % (lens.sum() / 16000.0, steps, impl, cores)},
        'e2e': {'value': value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from openeat_b200 import planner
    from openeat_b200._lib import OE_WAV_I16
    from openeat_b200.cmvn import all_reduce_stats
    from openeat_b200.dataset import PrefetchingCollator, _plan_batch, audio_collate_func
    from openeat_b200.frontend import aligned_offsets, default_frontend
    from openeat_b200.sharding import bind_to_gpu_numa_node, dynamic_batches, shard_by_length

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    t_start = time.perf_counter()

    def mark(what):                              # progress on stderr (one line per leg and rank): a stuck leg names itself
        print('[bench rank %d +%.1fs] %s' % (rank, time.perf_counter() - t_start, what), file=sys.stderr, flush=True)

    if os.environ.get('OE_BENCH_WATCHDOG'):      # dump every thread's stack and exit if the run takes longer than this many seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['OE_BENCH_WATCHDOG']), exit=True)
    cpus_before = os.sched_getaffinity(0)
    binding = bind_to_gpu_numa_node(local_rank)
    fe = default_frontend(80, 16000, dev)
    lens, speeds = workload(rank)
    host_pool, offs = synth_pool_host(lens, rank, POOL)
    dev_pool = [h.to(dev) for h in host_pool]
    keys = ['utt%d' % i for i in range(BATCH)]
    labels = [[1, 2, 3]] * BATCH
    audio_s = float(lens.sum()) / 16000.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))

    def job_total(x):
        """sum of a per-rank scalar over all ranks"""
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    job_audio_s = job_total(audio_s)          # audio seconds per step over ALL ranks (each rank has its own lengths)
    # synthetic global CMVN (any finite vectors exercise the same arithmetic)
    mean = torch.linspace(8.0, 12.0, 80, device=dev)
    istd = torch.linspace(0.4, 0.6, 80, device=dev)
    stats = torch.zeros(161, dtype=torch.float64, device=dev)

    # ---- value leg: every pool entry is planned (host RNG) and prepared (oe_batch_prepare) once ----
    random.seed(4242 + rank)
    preps, sp_frames = [], []
    for _ in range(POOL):
        plan = _plan_batch(keys, labels, lens, [16000] * BATCH, speeds, CONF)
        _, tm, fm = planner.plan_augment(plan.frames, 80, None, AUG)
        preps.append(fe.prepare(OE_WAV_I16, offs[plan.src], lens[plan.src], layout='padded', normalization=True,
                                tmask=tm, fmask=fm, cmvn=(mean, istd), cmvn_on_padding=True, stats=stats,
                                speed_ratios=plan.stage2))
        sp_frames.append(float(plan.frames.sum()))

    def step_resident(i):
        return fe.run(preps[i % POOL], dev_pool[i % POOL])

    collate = audio_collate_func(data_type='wav', feature_extraction_conf=CONF, normalization=True, spec_aug=True,
                                 spec_aug_conf=AUG, global_cmvn=(mean, istd), cmvn_stats=stats)

    def host_batches():
        i = 0
        while True:
            yield (host_pool[i % POOL], offs, lens, keys, labels, speeds)
            i += 1

    pipe = PrefetchingCollator(collate, host_batches())
    host_pad = os.environ.get('OE_BENCH_HOST_PAD') == '1'       # developer A/B: real rows over PCIe + padding on the host
    pipe_host = PrefetchingCollator(collate, host_batches(), to_host=True, host_pad=host_pad)
    d2h = {'n': torch.empty(BATCH, dtype=torch.int32).pin_memory(),
           's': torch.empty(161, dtype=torch.float64).pin_memory()}
    feat_bytes = [0]

    def step_e2e(i):
        """Public API from pinned host memory: H2D of this step's PCM (PrefetchingCollator: on a side stream,
        one batch ahead), all kernels, and a D2H read of the step's result (frame counts + the running CMVN
        statistics; the features stay on the GPU for the model)."""
        _, out = next(pipe)
        d2h['n'].copy_(out['features_length'], non_blocking=True)
        d2h['s'].copy_(stats, non_blocking=True)

    def step_e2e_host(i):
        """The reference boundary: the padded feature tensor comes back to (pinned) host memory as well."""
        _, out = next(pipe_host)
        if host_pad:
            feat_bytes[0] = int(out['features_length'].sum().item()) * out['features'].shape[-1] * 4     # the rows that were copied
        else:
            feat_bytes[0] = out['features'].numel() * 4
        d2h['s'].copy_(stats, non_blocking=True)

    def timed(fn, steps, warmup, with_allreduce):
        for i in range(warmup):
            fn(i)
        if with_allreduce and warmup:
            all_reduce_stats(stats)              # the collective is warmed up like everything else
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = fe.launches
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if with_allreduce:
            all_reduce_stats(stats)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), fe.launches - l0

    def timed_repeated(fn, with_allreduce, min_seconds, steps=None):
        """Times EXACTLY `steps` steps (barrier + sync on both sides); the K-step region is repeated until
        `min_seconds` have passed so that nvidia-smi (100 ms period) sees the clocks under this very load; the
        median repeat is reported."""
        steps = steps or args.steps
        runs, launches, t0 = [], 0, time.perf_counter()
        while True:
            ms, launches = timed(fn, steps, warmup if not runs else 0, with_allreduce)
            runs.append(ms)
            go = torch.tensor([1.0 if time.perf_counter() - t0 < min_seconds and len(runs) < 400 else 0.0], device=dev)
            if world > 1:
                dist.broadcast(go, 0)
            if go.item() == 0.0:
                break
        return float(np.median(runs)), launches, len(runs)

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    mark('device-resident')
    ms_total, launches, reps = timed_repeated(step_resident, world > 1, 1.5)
    clocks = sampler.stop() if rank == 0 else None
    value = job_audio_s * args.steps / (ms_total * 1e-3)
    stats.zero_()
    random.seed(99 + rank)
    mark('e2e packed')
    ms_e2e, _, _ = timed_repeated(step_e2e, world > 1, 0.5)
    e2e_value = job_audio_s * args.steps / (ms_e2e * 1e-3)
    mark('e2e to_host')
    ms_e2e_host, _, _ = timed_repeated(step_e2e_host, world > 1, 0.5)
    e2e_host_value = job_audio_s * args.steps / (ms_e2e_host * 1e-3)

    # ---- the same from WAV FILES (tmpfs): native ingest (header parse + multi-threaded pread into a pinned ring) on a
    # background thread -> H2D -> kernels; the reference reads its wavs inside DataLoader workers (train.py:110-116) ----
    import shutil
    import tempfile
    import wave
    from openeat_b200.ingest import ingest_batches
    wav_dir = tempfile.mkdtemp(prefix='oe_bench_%d_' % rank, dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    file_batches = []
    for bi in range(POOL):
        pcm = host_pool[bi].numpy()
        items = []
        for u in range(BATCH):
            path = os.path.join(wav_dir, 'b%d_u%d.wav' % (bi, u))
            with wave.open(path, 'wb') as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(16000)
                w.writeframes(pcm[offs[u]:offs[u] + lens[u]].tobytes())
            items.append((keys[u], path, labels[u], speeds[u]))
        file_batches.append(items)

    files_done = []

    def file_items():
        i = 0
        while not files_done:
            yield file_batches[i % POOL]
            i += 1

    # reader threads: every rank takes its share of the host cores (N ranks x all cores would oversubscribe the box)
    from openeat_b200.ingest import NativeIngest
    n_readers = max(2, len(os.sched_getaffinity(0)) // world)
    pipe_files = PrefetchingCollator(collate, ingest_batches(file_items(), ingest=NativeIngest(threads=n_readers, ring=6), depth=3))

    def step_e2e_files(i):
        _, out = next(pipe_files)
        d2h['n'].copy_(out['features_length'], non_blocking=True)
        d2h['s'].copy_(stats, non_blocking=True)

    def prewarm(step, n):
        # the pinned ring slots of a file-based leg are allocated (cudaHostAlloc, milliseconds each) the first time the ring
        # reaches them: walk the whole ring once before anything is timed, or the first timed repeat measures the allocator
        for i in range(n):
            step(i)
        torch.cuda.synchronize()

    mark('e2e from wav files')
    prewarm(step_e2e_files, 10)
    ms_e2e_files, _, _ = timed_repeated(step_e2e_files, world > 1, 0.5)
    e2e_files_value = job_audio_s * args.steps / (ms_e2e_files * 1e-3)
    files_done.append(True)                     # the background reader stops; the files go at exit
    import atexit
    atexit.register(shutil.rmtree, wav_dir, ignore_errors=True)

    # ---- the same from FLAC FILES (the LibriSpeech recipe's corpus format): the compressed bytes cross PCIe and the GPU
    # decodes them (oe_flac_pack -> oe_flac_decode_batch -> the same kernels).  White noise does not compress, so this leg
    # has its own corpus of the same shape: low-pass noise under a syllable-rate envelope whose FLAC stream is ~0.53 of the
    # PCM (LibriSpeech's own ratio), written by the library's encoder.  Next to it: the same files decoded on the host by the
    # reader threads (what a libsox-style loader does), same pipeline behind. ----
    import ctypes
    from openeat_b200 import _lib as oe_lib
    from openeat_b200.ingest import FlacGpuIngest, flac_gpu_batches
    lib = oe_lib.load()
    FLAC_POOL = 4
    flac_batches, flac_bytes, pcm_bytes = [], 0, 0
    rng_f = np.random.default_rng(1234 + rank)
    enc_out = np.empty(2 * int(max(lens)) + 65536, dtype=np.uint8)
    for bi in range(FLAC_POOL):
        total_n = int(sum(lens)) + 64
        x = rng_f.normal(0.0, 1.0, total_n).astype(np.float32)
        for _ in range(3):
            x = np.convolve(x, np.array([0.25, 0.5, 0.25], dtype=np.float32), mode='same')
        x /= x.std()
        env = np.clip(np.sin(2 * np.pi * np.arange(total_n, dtype=np.float32) / 5000.0), 0.02, None)
        sig = np.clip(np.round(1500.0 * x * env), -32768, 32767).astype(np.int16)
        items, at = [], 0
        for u in range(BATCH):
            seg = np.ascontiguousarray(sig[at:at + lens[u]])
            at += int(lens[u])
            nb = ctypes.c_int64()
            oe_lib.check(lib.oe_flac_encode(seg.ctypes.data, seg.size, 16000, 4096, 3, enc_out.ctypes.data, enc_out.size, ctypes.byref(nb)))
            path = os.path.join(wav_dir, 'b%d_u%d.flac' % (bi, u))
            with open(path, 'wb') as f:
                f.write(enc_out[:nb.value].tobytes())
            flac_bytes += nb.value
            pcm_bytes += 2 * seg.size
            items.append((keys[u], path, labels[u], speeds[u]))
        flac_batches.append(items)
    flac_done = []

    def flac_items():
        i = 0
        while not flac_done:
            yield flac_batches[i % FLAC_POOL]
            i += 1

    # half the cores: the pack is short (25 MB), and reader threads that outnumber the free cores slow the collate thread down
    # (collate_packed 1.03 ms per step with 16 reader threads on the 16-core box, 0.64 ms with 8: tools/flac_pipeline_profile.py)
    n_flac_readers = max(2, n_readers // 2)
    pipe_flac = PrefetchingCollator(collate, flac_gpu_batches(flac_items(), depth=3, workers=2, threads=n_flac_readers))

    def step_e2e_flac(i):
        _, out = next(pipe_flac)
        d2h['n'].copy_(out['features_length'], non_blocking=True)
        d2h['s'].copy_(stats, non_blocking=True)

    mark('e2e from flac files (GPU decode)')
    prewarm(step_e2e_flac, 18)
    ms_e2e_flac, _, _ = timed_repeated(step_e2e_flac, world > 1, 0.5)
    e2e_flac_value = job_audio_s * args.steps / (ms_e2e_flac * 1e-3)
    flac_done.append(True)
    g_tmp = FlacGpuIngest(threads=n_readers, ring=2)
    fb_tmp = g_tmp.pack([x[1] for x in flac_batches[0]], report=False)
    flac_h2d_bytes, flac_frames = int(fb_tmp.h2d_bytes), int(fb_tmp.n_frames)
    del fb_tmp, g_tmp
    flac_host_done = []

    def flac_host_items():
        i = 0
        while not flac_host_done:
            yield flac_batches[i % FLAC_POOL]
            i += 1

    pipe_flac_host = PrefetchingCollator(collate, ingest_batches(flac_host_items(), ingest=NativeIngest(threads=n_readers, ring=6), depth=3))

    def step_e2e_flac_host(i):
        _, out = next(pipe_flac_host)
        d2h['n'].copy_(out['features_length'], non_blocking=True)
        d2h['s'].copy_(stats, non_blocking=True)

    flac_host_steps = max(2, min(args.steps, 5))
    mark('e2e from flac files (host decode)')
    prewarm(step_e2e_flac_host, 8)
    ms_e2e_flac_host, _, _ = timed_repeated(step_e2e_flac_host, world > 1, 0.3, steps=flac_host_steps)
    e2e_flac_host_value = job_audio_s * flac_host_steps / (ms_e2e_flac_host * 1e-3)
    flac_host_done.append(True)

    # ---- the box's own H2D ceiling: plain cudaMemcpyAsync from pinned memory, every rank at once ----
    h2d = int(host_pool[0].numel() * 2)
    sink = torch.empty_like(dev_pool[0])

    def step_copy(i):
        sink.copy_(host_pool[i % POOL], non_blocking=True)

    mark('h2d ceiling')
    ms_copy, _, _ = timed_repeated(step_copy, False, 0.3)
    h2d_job = job_total(h2d)                                                     # bytes per step over all ranks
    h2d_ceiling = h2d_job * args.steps / (ms_copy * 1e-3) / 1e9                  # GB/s over all ranks

    mark('other configs')
    # ---- the other BASELINE configurations, device-resident ----
    def run_config(name, make):
        """make() -> (list of (prepared batch, device wav), audio seconds per pass on this rank, algorithmic bytes per
        pass, needs the stats all-reduce).  One timed repeat = one pass over the list."""
        items, secs, alg, collective = make()
        n_items = len(items)

        def one_pass(i):
            for p, w in items:
                fe.run(p, w)

        ms, ln, _ = timed_repeated(one_pass, collective and world > 1, 0.4, steps=3)
        ms /= 3.0
        tot_s, tot_b = job_total(secs), job_total(alg)
        return {'workload': name, 'audio_s_per_s': tot_s / (ms * 1e-3), 'ms_per_pass': ms, 'batches_per_pass_per_gpu': n_items,
                'gpu_launches_per_batch': ln / 3.0 / max(1, n_items),
                'alg_gb_s_per_gpu': tot_b / world / (ms * 1e-3) / 1e9, 'hbm_frac': tot_b / world / (ms * 1e-3) / 1e9 / peak}

    def frames_of(n):
        return fe.num_frames_array(n)

    def make_cfg1():
        # configs[0] shape: 1 000 x 5.000 s, static batches of 16, per-utt norm (recipe default) + global CMVN
        n = np.full(16, 80000, dtype=np.int32)
        o, total = aligned_offsets(n)
        nb = 63                                                 # 1 008 utterances, 161 MB of PCM > L2
        wav = synth_dev(total * nb, dev, 2001 + rank)
        p = fe.prepare(OE_WAV_I16, o, n, layout='padded', normalization=True, cmvn=(mean, istd), cmvn_on_padding=True)
        items = [(p, wav[i * total:(i + 1) * total]) for i in range(nb)]
        fr = float(frames_of(n).sum())
        return items, nb * 16 * 5.0, nb * (2.0 * float(n.sum()) + 320.0 * fr), False

    def make_cfg3():
        # configs[2] shape: ONE global list of U[1,35] s utterances (seed 1004; a sample of the 1 000 h set: 10 h per GPU,
        # so that the work per GPU is the same at every N like the rest of this line -- the full set would give every GPU
        # 125 h at N = 8, a 10 h total would leave each of eight GPUs 0.7 ms of work), cut by shard_by_length, sorted +
        # length-bucketed like AudioDataset(sort=True, batch_type='dynamic'), fbank + per-utt norm + padded output +
        # CMVN-statistics accumulation, one all-reduce per pass
        rng = np.random.default_rng(1004)
        all_lens = np.round(rng.uniform(1.0, 35.0, 2000 * world) * 16000).astype(np.int64)
        mine = shard_by_length(all_lens, world)[rank]
        n = all_lens[mine].astype(np.int32)
        fr = frames_of(n)
        items, secs, alg = [], 0.0, 0.0
        for b in dynamic_batches([int(v) for v in fr], 150000, sort=True):
            nb_ = n[b][::-1].copy()                             # longest first, like the collate's sort
            o, total = aligned_offsets(nb_)
            wav = synth_dev(total, dev, 3001 + 17 * len(items) + rank)
            p = fe.prepare(OE_WAV_I16, o, nb_, layout='padded', normalization=True, stats=stats)
            items.append((p, wav))
            secs += float(nb_.sum()) / 16000.0
            alg += 2.0 * float(nb_.sum()) + 320.0 * float(frames_of(nb_).sum())
        return items, secs, alg, True

    def make_cfg4():
        # configs[3] shape: 256 x U[0.5,3] s, 80-mel + per-utt norm + spec_sub (3, 30): launch / framing overhead
        rng = np.random.default_rng(1005 + 7919 * rank)
        n = np.sort(np.round(rng.uniform(0.5, 3.0, BATCH) * 16000).astype(np.int32))[::-1].copy()
        o, total = aligned_offsets(n)
        nb = 20                                                 # 287 MB of PCM > L2
        wav = synth_dev(total * nb, dev, 4001 + rank)
        fr = frames_of(n)
        random.seed(777 + rank)
        items = []
        for i in range(nb):
            fmap, _, _ = planner.plan_augment(fr, 80, SUB, None)
            p = fe.prepare(OE_WAV_I16, o, n, layout='padded', normalization=True, frame_maps=fmap)
            items.append((p, wav[i * total:(i + 1) * total]))
        return items, nb * float(n.sum()) / 16000.0, nb * (2.0 * float(n.sum()) + 320.0 * float(fr.sum())), False

    def make_cfg5():
        # configs[4] shape: 20-minute streams (19.2 M samples) fed as 7 500 windows of 16 frames (2 800 samples, 240
        # overlap: windows are views into the stream), 80-mel + global CMVN, ragged output == whole-stream fbank
        nwin, hop, wl = 7500, 2560, 2800
        total = hop * (nwin - 1) + wl + 8
        nb = 4                                                  # 154 MB of PCM > L2
        wav = synth_dev(total * nb, dev, 5001 + rank)
        o = np.arange(nwin, dtype=np.int64) * hop
        n = np.full(nwin, wl, dtype=np.int32)
        p = fe.prepare(OE_WAV_I16, o, n, layout='ragged', cmvn=(mean, istd))
        items = [(p, wav[i * total:(i + 1) * total]) for i in range(nb)]
        return items, nb * nwin * 16 * 0.01, nb * (2.0 * (hop * nwin + 240) + 320.0 * 16 * nwin), False

    configs = {}
    for key, name, mk in (('config1_shape', 'configs[0] shape on the GPU: 1 008 x 5 s, static batch 16, per-utt norm + global CMVN', make_cfg1),
                          ('config3', 'configs[2]: LibriSpeech shape, %d x U[1,35] s (10 h per GPU sample), one list cut by shard_by_length, '
                                      'dynamic batches of <= 150 k frames, per-utt norm + CMVN statistics, one all-reduce per pass' % (2000 * world), make_cfg3),
                          ('config4', 'configs[3]: ASRU shape, 256 x U[0.5,3] s, per-utt norm + spec_sub(3,30)', make_cfg4),
                          ('config5', 'configs[4]: 20-minute streams as 7 500 x 16-frame windows, global CMVN, ragged output', make_cfg5)):
        stats.zero_()
        configs[key] = run_config(name, mk)
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (k2::oe_fbank2_kernel) ----
    roof = None
    cpu = None
    if rank == 0:
        frames = fe.num_frames_array(lens)
        alg_plain = 2.0 * float(lens.sum()) + 4.0 * 80 * float(frames.sum())       # int16 in, fp32 out
        out = torch.empty((int(frames.sum()), 80), device=dev)
        plain = fe.prepare(OE_WAV_I16, offs, lens, layout='ragged')
        for i in range(3):
            fe.run(plain, dev_pool[i % POOL], out=out)
        torch.cuda.synchronize()
        durs = []
        for rep in range(5):                         # 5 x 16 back-to-back launches: the GPU never waits for the host
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(16):
                fe.run(plain, dev_pool[i % POOL], out=out)
            b.record()
            torch.cuda.synchronize()
            durs.append(a.elapsed_time(b) / 16.0)
        plain_ms = float(np.median(durs))
        # the kernel as it runs INSIDE the timed step (fused speed perturb on 2/3 of the utterances, tile + global
        # statistics, raw rows straight into the padded tensor): the library brackets it with CUDA events on the stream
        fe.set_kernel_timing(True)
        in_step, whole = [], []
        for i in range(12):
            step_resident(i)
            in_step.append(fe.fbank_kernel_ms())
            whole.append(fe.step_ms())
        fe.set_kernel_timing(False)
        step_ms = float(np.median(in_step[2:]))
        alg_step = 2.0 * float(lens.sum()) + 4.0 * 80 * float(np.mean(sp_frames))   # frames after the speed perturb
        traffic, traffic_note = None, None
        try:
            t = json.load(open(os.path.join(ROOT, 'profiles', 'r02_step_traffic.json')))
            traffic, traffic_note = t.get('fbank_dram_bytes_per_launch'), t.get('how')
        except Exception:
            pass
        achieved = alg_step / (step_ms * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_how': traffic_note,
                'kernel': 'k2::oe_fbank2_kernel<int16, fused speed perturb> + tile / global statistics, as launched inside the timed step',
                'launch_ms': step_ms, 'alg_bytes_per_launch': alg_step,
                'share_of_step': step_ms / (ms_total / args.steps),
                'how': 'CUDA events recorded by the library around the kernel launch on the stream, median of 10 steps '
                       '(the events end the programmatic overlap with the neighbouring kernels, so the sum of the parts '
                       'exceeds the step)',
                'step_frac': alg_step / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                'launch_sequence_ms_with_events': float(np.median(whole[2:])),
                'plain_instantiation': {'kernel': 'k2::oe_fbank2_kernel<int16> alone (ragged output, no resampler, no statistics)',
                                        'launch_ms': plain_ms, 'alg_bytes_per_launch': alg_plain,
                                        'achieved': alg_plain / (plain_ms * 1e-3) / 1e9,
                                        'frac': alg_plain / (plain_ms * 1e-3) / 1e9 / peak},
                'peak_source': 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s',
                'note': 'not HBM-bound: FMA pipe, issue slots and the shared-memory pipe are each ~50 % busy '
                        '(10.6 k FP32 lane-ops and 106 smem wavefronts per frame); see DESIGN.md section 4.1'}
        if world == 1:
            os.sched_setaffinity(0, cpus_before)         # the CPU baseline gets every core of the box, not the GPU-local ones
            v, cores, cms, impl = cpu_reference_throughput(lens, speeds, 3, 1)
            cpu = {'value': v, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
                   'sample': '3 x one 256-utterance batch (%.0f audio-s each); oracle collate port calling %s, torchaudio.functional.speed, '
                             'numpy norm/spec_aug on %d processes x 1 thread' % (audio_s, impl, cores)}

    if rank == 0:
        line = {
            'metric': 'fbank_audio_seconds_per_second', 'value': value, 'unit': 'audio-s/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'audio_s_per_step_per_gpu': job_audio_s / world,
                       'l2': 'inputs cycle through %d distinct batches (%.0f MB int16 per GPU) > 126 MB L2' %
                             (POOL, POOL * h2d / 1e6),
                       'timing': 'median of %d back-to-back repeats of the %d-step timed region (each bracketed by '
                                 'barrier + synchronize; repeats only lengthen the window nvidia-smi samples)' % (reps, args.steps),
                       'value_leg': 'batches prepared once per pool entry (host RNG plan + oe_batch_prepare); a step = '
                                    'oe_fbank_run: 4 kernels (metadata fetch from mapped pinned memory, descriptors, fbank, in-place completion)',
                       'parallelism': 'utterance sharding, dp%d; one 161 x f64 NCCL all-reduce closes the timed region'
                                      % world if world > 1 else 'single GPU',
                       'host_binding': binding or 'none'},
            'e2e': {'value': e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': BATCH * 4 + 161 * 8, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'openeat_b200.dataset.PrefetchingCollator over audio_collate_func.collate_packed (pinned int16 '
                           '-> GPU features; H2D of batch i+1 overlaps batch i; frame counts + CMVN stats read back '
                           'every step; the feature tensor stays on the GPU for the model)',
                    'to_host': {'value': e2e_host_value, 'unit': 'audio-s/s', 'ms_per_step': ms_e2e_host / args.steps,
                                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(feat_bytes[0]) + BATCH * 8 + 161 * 8,
                                'api': 'the same with to_host=True: the padded feature tensor returns to pinned host memory '
                                       'on a third stream (the reference boundary: audio_collate_func returns CPU tensors)'
                                       + ('; OE_BENCH_HOST_PAD=1: only the real rows cross PCIe, oe_host_pad_rows completes '
                                          'the padded tensor on the host' if host_pad else '')},
                    'from_wav_files': {'value': e2e_files_value, 'unit': 'audio-s/s', 'ms_per_step': ms_e2e_files / args.steps,
                                       'api': 'openeat_b200.ingest.ingest_batches (native RIFF parse + multi-threaded pread of %d wav '
                                              'files per step from tmpfs into a pinned ring, background thread, %d reader threads per rank) -> '
                                              'PrefetchingCollator -> the same kernels' % (BATCH, n_readers),
                                       'frac_of_packed_e2e': e2e_files_value / e2e_value},
                    'from_flac_files': {'value': e2e_flac_value, 'unit': 'audio-s/s', 'ms_per_step': ms_e2e_flac / args.steps,
                                        'h2d_bytes_per_step': flac_h2d_bytes, 'flac_over_pcm_bytes': flac_bytes / pcm_bytes,
                                        'frames_per_step': flac_frames,
                                        'corpus': 'same utterance lengths, speech-like synthetic signal (low-pass noise under a '
                                                  'syllable-rate envelope), oe_flac_encode block 4096',
                                        'api': 'openeat_b200.ingest.flac_gpu_batches (oe_flac_submit / oe_flac_wait: pread of %d FLAC files per step + frame '
                                               'index on native driver threads, nothing decoded on the host) -> PrefetchingCollator: compressed bytes over PCIe, '
                                               'oe_flac_decode_batch (one kernel, end-of-frame + CRC-16 checks) -> the same kernels' % BATCH,
                                        'frac_of_packed_e2e': e2e_flac_value / e2e_value,
                                        'host_decode': {'value': e2e_flac_host_value, 'unit': 'audio-s/s',
                                                        'ms_per_step': ms_e2e_flac_host / flac_host_steps,
                                                        'api': 'the same files through ingest_batches: %d reader threads decode FLAC on the '
                                                               'host (oe_flac.h), PCM over PCIe' % n_readers},
                                        'gpu_over_host_decode': e2e_flac_value / e2e_flac_host_value},
                    'h2d_ceiling_gbs': h2d_ceiling,
                    'h2d_achieved_gbs': h2d_job * args.steps / (ms_e2e * 1e-3) / 1e9,
                    'frac_of_h2d_ceiling': (h2d_job * args.steps / (ms_e2e * 1e-3) / 1e9) / h2d_ceiling},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'configs': configs,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line), file=JSON_OUT, flush=True)


def main():
    # stdout carries exactly ONE JSON line: anything the mirrored reference code prints (e.g. dataset.py:183's
    # 'normalize feature ...') goes to stderr
    # (and so does anything native code writes to fd 1, e.g. NCCL's version banner under NCCL_DEBUG=VERSION)
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', init_method='env://', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
