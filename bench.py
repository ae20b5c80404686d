"""bench.py -- train time-steps/sec of Composer LSTM-MultiNADE on synthetic piano-rolls (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (CPU restatement of the reference's TF1 graph)

A "step" is one full training step (fwd + bwd + allreduce + clip + Adam) over the global batch [B,T,84,5];
strong scaling: the global batch is fixed and sharded over ranks. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'C5': dict(B=2048, T=256, mode='composer', name='composer-lstm-multinade [2048,256,84,5] (BASELINE configs[4])'),
    'C2': dict(B=256, T=128, mode='composer', name='composer-lstm-multinade [256,128,84,5] (BASELINE configs[1])'),
    # the other BASELINE configs: parity-test cases (tools/config_check.py), timed here on request with the same contract line
    'C1': dict(B=64, T=64, mode='jamming', name='jamming pass + lstm-nade [64,64,84,5] (BASELINE configs[0])',
               kw=dict(encoder='Pass', generator='NADE')),
    'C3': dict(B=512, T=128, mode='joint', name='joint dbn + lstm-rbm k=10 [512,128,84,5] (BASELINE configs[2])',
               kw=dict(encoder='DBN', encoder_hidden=[168, 84], generator='RBM')),
    'C4': dict(B=1024, T=256, mode='feedback-rnn',
               name='feedback-rnn dbn + lstm-nade [1024,256,84,5] (BASELINE configs[3])',
               kw=dict(encoder='DBN', encoder_hidden=[168, 84], generator='NADE', num_hidden_rnn=(256, 256),
                       feedback=[256, 128])),
}
D, M, H, RNN = 84, 5, 256, (512, 256)
# algorithmic forward flops per time-step (one (b,t) row, all 5 tracks); fwd+bwd = 3x by the 1:2 convention
DENSE_FLOPS_FWD = 6_260_736          # LSTM [932x2048 + 768x1024] + Dense [256x1700], 2 flops per MAC
NADE_FLOPS_FWD = 9_139_200
# split of DENSE_FLOPS_FWD: batched GEMMs (input projections + Dense) vs the sequential recurrence h.Wh
BATCHED_FLOPS_FWD = 2 * (420 * 2048 + 512 * 1024 + 256 * 1700)   # 3 639 296
RECUR_FLOPS_FWD = 2 * (512 * 2048 + 256 * 1024)                  # 2 621 440
# what one training step really asks of mnn_gemm_tc per time-step: forward input projections + Dense, their data
# gradients (none for layer 0: the inputs are data) and ALL weight gradients (the recurrent ones are batched GEMMs too)
GEMM_TC_FLOPS_STEP = (BATCHED_FLOPS_FWD                                   # forward
                      + 2 * (256 * 1700 + 512 * 1024)                     # dout = dfc.K^T, d_in2 = dG2.W2x^T
                      + 2 * (932 * 2048 + 768 * 1024 + 256 * 1700))       # dW1, dW2 (x and h rows), dK
# NADE in the segment form the kernels run (exact for any input): decode dots D*H MACs per track, forward; the backward
# needs the same dots twice (d h and d W_dec)
NADE_SEG_FLOPS_FWD = 2 * M * D * H
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # CUDA-core FMA peak of a B200 at 1965 MHz (SIMT kernels)


def profiled_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/*traffic.json), or None."""
    for name in ('r2_final_traffic.json', 'r2_traffic.json', 'r1_traffic.json'):
        p = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(p):
            return json.load(open(p)).get(kernel)
    return None


def profiled_step_traffic():
    """Sum of dram bytes over every kernel of one training step from the same capture (C5 shapes), with its source."""
    for name in ('r2_final_traffic.json', 'r2_traffic.json', 'r1_traffic.json'):
        p = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(p):
            d = json.load(open(p))
            tot = sum(v.get('dram_bytes_per_step', 0.0) for v in d.get('detail', {}).values())
            return {'dram_bytes_per_step': tot, 'source': 'profiles/' + name}
    return None


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(B, T, steps, warmup, threads=None):
    """The reference's training step restated op for op on torch-CPU (oracle/torch_ref.py): unrolled 84-iteration
    NADE loop per track, per-step LSTM loop, autograd, clip 5, TF-Adam. Returns (time-steps/s, cores, loss)."""
    import torch
    from oracle import np_oracle as O
    from oracle import torch_ref as R
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = R.to_torch(O.init_composer_params(D, M, H, RNN, seed=23), torch.float32, requires_grad=True)
    opt = R.TFAdam(R.flat_params(params), lr=0.01)
    x = torch.from_numpy(O.synthetic_pianoroll(B, T, D, M, seed=23))
    rng = np.random.default_rng(0)
    times, loss = [], None
    for it in range(warmup + steps):
        u = [torch.from_numpy(rng.random((T, B, r), dtype=np.float32)) for r in RNN]
        t0 = time.perf_counter()
        loss, _ = R.composer_train_step(x, params, opt, keep=0.9, u_drop=u)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return B * T / float(np.mean(times)), cores, loss, float(np.mean(times))


def config_of(wl, world):
    """The `config` object of the JSON line: identical for both arms (the reference arm times a bounded SAMPLE of this
    workload, described in its cpu_baseline.sample)."""
    return {'workload': wl['name'], 'global_batch': wl['B'], 'time_steps': wl['T'], 'per_gpu_batch': wl['B'] // world,
            'keep_prob': 0.9, 'parallelism': f'dp{world}', 'input_dtype': 'uint8 piano-rolls',
            'l2_policy': 'per-step working set (~20 GB of activations at C5) exceeds the 126 MB L2'}


def cpu_jamming_epoch(n_songs=64, T=64, batch=32, threads=None):
    """BASELINE configs[0] on the host cores: Jamming (5 independent LSTM-NADE generators, one joint clip + Adam,
    multinn_jamming.py:235-243) on synthetic [64,64,84,5] piano-rolls, ONE epoch with the reference's default batch of 32
    (default_config.yaml:33) after one untimed warm-up step. Returns (time-steps/s, cores, seconds per epoch, loss)."""
    import torch
    from oracle import np_oracle as O
    from oracle import torch_ref as R
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    plist = [R.to_torch(O.init_rnn_nade_params(D, D, H, RNN, seed=23 + m), torch.float32, requires_grad=True)
             for m in range(M)]
    leaves = [p for pl in plist for p in R.flat_params(pl)]
    opt = R.TFAdam(leaves, lr=0.01)
    X = torch.from_numpy(O.synthetic_pianoroll(n_songs, T, D, M, seed=23))
    rng = np.random.default_rng(0)

    def step(x):
        u = [[torch.from_numpy(rng.random((T, x.shape[0], r), dtype=np.float32)) for r in RNN] for _ in range(M)]
        loss, _ = R.jamming_loss(x, plist, keep=0.9, u_drop=u)
        grads = torch.autograd.grad(loss, leaves)
        opt.step(grads)
        return float(loss.detach())

    step(X[:batch])
    t0 = time.perf_counter()
    loss = None
    for i in range(0, n_songs, batch):
        loss = step(X[i:i + batch])
    sec = time.perf_counter() - t0
    return n_songs * T / sec, cores, sec, loss


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    base = {'impl': 'reference', 'unit': 'time-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_of(wl, max(1, args.gpus)), 'gpu_launches': 0}
    if wl['mode'] == 'jamming':
        v, cores, sec, loss = cpu_jamming_epoch(wl['B'], wl['T'])
        sample = (f'the whole config: 1 epoch = {wl["B"]} songs x {wl["T"]} steps in batches of 32 ({sec:.2f} s), torch-CPU '
                  'fp32 restatement of the TF1 graph, keep_prob 0.9')
        print(json.dumps(dict(base, metric='train time-steps/sec (Jamming LSTM-NADE)', value=v, ms_per_step=sec * 1e3 / 2,
                              cpu_baseline={'value': v, 'unit': 'time-steps/s', 'cores': cores, 'kind': 'port',
                                            'sample': sample},
                              e2e={'value': v, 'unit': 'time-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                              final_loss=loss)))
        return
    if wl['mode'] != 'composer':
        print(json.dumps({'impl': 'reference', 'unavailable': f'no CPU restatement of a whole {wl["mode"]} training step '
                                                              '(oracle/ holds its pieces; the headline workload is C5)'}))
        return
    sT = wl['T']
    steps = max(1, args.steps)
    # the slice batch of the bounded sample is picked by a sweep (the 84 x 5 stored [N,H] autograd tensors make the port
    # memory-bound and batch-sensitive): one step at 32 / 64 / 128 rows, the fastest in time-steps/s is timed
    sweep = {}
    for sB in (32, 64, 128):
        if sB > args.cpu_batch:
            break
        sweep[sB] = cpu_port_run(sB, sT, 1, 0)[0]
        if sB * sT / sweep[sB] * (steps + 1) > args.cpu_seconds:       # the next size would not fit the time bound
            break
    sB = max(sweep, key=sweep.get)
    cap = int(args.cpu_seconds * sweep[sB] / ((steps + 1) * sT)) // 8 * 8
    sB = max(8, min(sB, cap))
    v, cores, loss, sec = cpu_port_run(sB, sT, steps, min(args.warmup, 1))
    sample = (f'[{sB},{sT},84,5] slice of the workload per step ({sec:.2f} s/step; slice batch picked by a sweep: '
              + ', '.join(f'{k} rows {x:.0f} ts/s' for k, x in sweep.items())
              + '), torch-CPU fp32 op-for-op restatement of the TF1 graph (TF 1.13.1 not installable), keep_prob 0.9')
    print(json.dumps(dict(base, metric='train time-steps/sec (Composer LSTM-MultiNADE)', value=v, ms_per_step=sec * 1e3,
                          cpu_baseline={'value': v, 'unit': 'time-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
                          e2e={'value': v, 'unit': 'time-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                          final_loss=loss)))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from multinn_b200 import _lib
    from multinn_b200.multinn import MultINN, default_config, default_params

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    wl = dict(WORKLOADS[args.workload])
    if args.batch:      # development aid: time one shard of a larger data-parallel job on one GPU (e.g. --batch 256 = dp8)
        wl['B'] = args.batch
        wl['name'] += f' [batch overridden to {args.batch}]'
    B, T = wl['B'], wl['T']
    assert B % world == 0
    Bl = B // world
    composer = wl['mode'] == 'composer'
    model = MultINN(default_config(), default_params(mode=wl['mode'], keep_prob=0.9, **wl.get('kw', {})), wl['mode'])
    step = model.train_generators('adam', 0.01)

    # synthetic Bernoulli(0.05) piano-rolls, two alternating host batches in pinned memory
    rng = np.random.default_rng(23 + rank)
    # kept as bytes like the reference's bool .npy piano-rolls (prepare_data.py:56), which it feeds to its float32
    # placeholder unchanged; the staging kernel widens them on the device
    hosts = [torch.from_numpy((rng.random((Bl, T, D, M)) < 0.05).astype(np.uint8)).pin_memory() for _ in range(2)]
    xdev = [h.cuda(non_blocking=True) for h in hosts]
    xbuf = torch.empty_like(xdev[0])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(n):
            fn(i)
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / n      # host time to ENQUEUE a step (no sync inside fn)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    losses = []
    resident = lambda i: losses.append(step(xdev[i & 1]))

    from multinn_b200.training import BatchPrefetcher
    pf = BatchPrefetcher()

    def e2e_step(i):
        # public input path: pinned host batch -> device (side stream, overlapping the previous step) -> step -> loss
        if pf.pending == 0:
            pf.put(hosts[i & 1])                          # H2D of this step's inputs from pinned memory
        x = pf.get()
        pf.put(hosts[(i + 1) & 1])                        # next step's inputs, copied while this step computes
        l = step(x)
        pf.release()
        losses.append(float(l))                           # D2H read of the step's loss

    for i in range(args.warmup):
        resident(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.lib.mnn_launch_count()
    ms = timed(resident, args.steps)
    host_enqueue_ms = host_ms[0]
    launches = (_lib.lib.mnn_launch_count() - n0)
    clk = clocks.stop() if rank == 0 else None
    final_loss = float(losses[-1])
    e2e_step(0)
    torch.cuda.synchronize()
    pf.__init__()        # nothing prefetched outside the timed region: its first step copies its own inputs
    ms_e2e = timed(e2e_step, args.steps)

    # ---- per-phase device times of one more step (CUDA events on the launching stream) -> roofline
    phases = profile_phases(model, xdev[0], args) if composer else {}
    # ---- XU (MUFU) pipe peak, measured here: MEASURED_PEAKS.json has no figure for the pipe that bounds the NADE sigmoids
    xu = None
    if rank == 0:
        from multinn_b200 import ops as _ops
        try:
            xu = {k: _ops.probe_mufu(k) for k in ('ex2', 'rcp', 'sigmoid')}
        except Exception as e:          # noqa: BLE001 - a diagnostic must not cost the bench line
            xu = {'error': repr(e)[:200]}

    # ---- autoregressive sampling (BASELINE configs[4]: ALL 512 steps from a 32-step intro, sample.py). After the bench's
    # training steps the model draws frames about as dense as its data (~4-5 %), the regime of any trained model; the
    # worst case for the segment-form sampler (~50 % dense frames, what random-init weights draw) is emulated by
    # shifting the decoder-bias columns of the Dense layer by +3 for a second timing
    sampling = None
    if not args.no_sampling and composer:
        S = 512
        intro = xdev[0][:, :32].contiguous()
        model.generate(intro, 4, seed=1)
        gen = model._model.generators[0]

        def time_generate(seed):
            nonlocal intro
            n0 = _lib.lib.mnn_launch_count()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = model.generate(intro, S, seed=seed)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / S * 1e3
            return {'batch': int(intro.shape[0]), 'us_per_generated_step': us,
                    'generated_time_steps_per_s': world * intro.shape[0] / (us * 1e-6),
                    'sample_density': float(out.mean()), 'launches_per_step': (_lib.lib.mnn_launch_count() - n0) / S}
        sparse_run = time_generate(2)
        bias = gen._fc_bias.data
        saved = bias.clone()
        bias[M * H:] += 3.0
        dense_run = time_generate(3)
        bias.copy_(saved)
        # the batch sample.py really runs (default_config.yaml: num_songs 3 x (16 + 8) intros = 72 rows): the one-launch
        # persistent kernel (mnn_generate_fused) against the per-step loop
        from multinn_b200 import ops as _gops
        small = {}
        intro_small = xdev[0][:72, :32].contiguous()
        saved_intro, saved_mode = intro, _gops.GENERATE_MODE
        try:
            intro = intro_small
            for mode_name in ('fused', 'steps'):
                _gops.GENERATE_MODE = mode_name
                model.generate(intro, 4, seed=1)
                small[mode_name] = time_generate(5)
        finally:
            intro, _gops.GENERATE_MODE = saved_intro, saved_mode
        sampling = {'batch_per_gpu': Bl, 'intro_steps': 32, 'timed_steps': S, 'as_trained': sparse_run,
                    'batch_72_fused_one_launch': small.get('fused'), 'batch_72_step_loop': small.get('steps'),
                    'decoder_bias_plus_3': dense_run,
                    'us_per_generated_step': sparse_run['us_per_generated_step'],
                    'generated_time_steps_per_s': sparse_run['generated_time_steps_per_s'],
                    'note': 'headline = the model as the bench left it (frames as dense as the data); decoder_bias_plus_3 = '
                            'dense frames, the worst case of the segment-form sampler'}

    # ---- size-independent property at the FULL bench size (after every timed region; never fatal): a sequence's per-row
    # NLL does not depend on what else is in the batch, although B = 2048 runs the pair kernels and B = 8 the 1-CTA ones
    fullsize = None
    if rank == 0 and world == 1 and composer:
        try:
            sub = 8
            full = model.evaluate(xdev[0])['nll'][:sub * T].clone()
            part = model.evaluate(xdev[0][:sub].contiguous())['nll']
            rel = float(((full - part).abs() / part.abs().clamp_min(1e-6)).max())
            fullsize = {'property': f'per-row NLL of the first {sub} sequences inside the [{Bl},{T},84,5] batch equals the '
                                    'same sequences evaluated alone (rows n = b*T + t)', 'rows': sub * T,
                        'max_rel_diff': rel, 'tolerance': 1e-4, 'ok': bool(rel < 1e-4)}
        except Exception as e:          # noqa: BLE001 - a diagnostic must not cost the bench line
            fullsize = {'error': repr(e)[:300]}

    # ---- RBM Gibbs chain of config C3's generator (84 x 256, k = 10): fused one-launch kernel vs the GEMM + half-step
    # path at generation and training row counts (tools/gibbs_bench.py; after every timed region; never fatal)
    gibbs = None
    if rank == 0 and world == 1 and not args.no_sampling and composer:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location('gibbs_bench', os.path.join(ROOT, 'tools', 'gibbs_bench.py'))
            gb = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(gb)
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                gibbs = gb.main()
        except Exception as e:          # noqa: BLE001
            gibbs = {'error': repr(e)[:300]}

    if rank == 0:
        pk = peaks()
        tps = B * T / (ms * 1e-3)
        n_rows = Bl * T
        gemm_ms = phases.get('gemm_ms', 0.0)
        gemm_tf = GEMM_TC_FLOPS_STEP * n_rows / (gemm_ms * 1e-3) / 1e12 if gemm_ms else 0.0
        n_gemm = max(1, phases.get('gemm_launches', 1))
        metric = {'composer': 'train time-steps/sec (Composer LSTM-MultiNADE)',
                  'jamming': 'train time-steps/sec (Jamming LSTM-NADE)',
                  'joint': 'train time-steps/sec (Joint DBN + LSTM-RBM, CD-k k=10)',
                  'feedback-rnn': 'train time-steps/sec (Feedback-RNN DBN + LSTM-NADE)'}[wl['mode']]
        table = kernel_table(phases, n_rows, pk, xu) if composer else []
        top = next((r for r in table if r['kernel'].startswith('nade_bwd')), None)
        step_traffic = profiled_step_traffic()
        if composer:
            roofline = {
                'bound': 'tensor',
                'kernel': 'mnn::tc::gemm_tc2_kernel / gemm_tc_kernel (tcgen05 kind::f16 on bf16 operand pairs, cta_group::2 pair tiles: '
                          'input projections, Dense, data- and weight-gradient GEMMs; largest kernel CLASS of the step)',
                'achieved': gemm_tf, 'peak': pk['tf_sust'], 'unit': 'TFLOP/s', 'frac': gemm_tf / pk['tf_sust'],
                'traffic': profiled_traffic('gemm_tc'), 'peak_source': pk['src'] + ' (cuBLAS bf16, sustained)',
                'launches_per_step': n_gemm, 'avg_launch_ms': gemm_ms / n_gemm,
                'algorithmic_flops_per_step': GEMM_TC_FLOPS_STEP * n_rows,
                'note': 'training-step GEMMs split every fp32 operand into a bf16 pair x1 + x2 and issue A1.B1 + A1.B2 + A2.B1 '
                        '(3 bf16 MMAs per algorithmic MAC, 2 with a binary A; ~2^-17 per product), so the ceiling against '
                        'the bf16 denominator is 1/3; measured bound: shared-memory bandwidth of the in-kernel split',
                # the single largest KERNEL of the step is the SIMT NADE backward: reported beside the class above
                'top_kernel': top,
                # HBM bytes of the whole step: profiled (ncu --set full, C5 shapes) against SURVEY 8(d)'s algorithmic figure
                'step_traffic': {'profiled': step_traffic, 'algorithmic_bytes_per_step': 55_000 * n_rows,
                                 'note': 'SURVEY 8(d): ~55 KB per time-step'},
                'phases_ms': phases}
        else:
            # the other BASELINE configs share the kernels profiled on C5; their line carries the step time only
            roofline = {'bound': 'tensor', 'kernel': 'same kernel classes as C5 (gemm_tc2 / lstm_tc / nade / rbm_gibbs); not '
                                                     'broken down for this workload', 'achieved': None, 'peak': pk['tf_sust'],
                        'unit': 'TFLOP/s', 'frac': None, 'traffic': None}
        out = {
            'metric': metric, 'value': tps, 'unit': 'time-steps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config_of(wl, world),
            'e2e': {'value': B * T / (ms_e2e * 1e-3), 'unit': 'time-steps/s', 'ms_per_step': ms_e2e,
                    'h2d_bytes_per_step': hosts[0].numel() * hosts[0].element_size(), 'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches),
            'host_enqueue_ms_per_step': host_enqueue_ms,
            'clocks': clk,
            'roofline': roofline,
            'kernels': table,
            'xu_peak': xu,
            'sampling': sampling,
            'final_loss': final_loss,
            'fullsize_check': fullsize,
            'rbm_gibbs_chain_84x256_k10': gibbs,
        }
        if world == 1 and not args.no_cpu and composer:
            v, cores, _, sec = cpu_port_run(min(args.cpu_batch, 64), T, 3, 1)
            out['cpu_baseline'] = {'value': v, 'unit': 'time-steps/s', 'cores': cores, 'kind': 'port',
                                   'sample': f'[{min(args.cpu_batch, 64)},{T},84,5] slice, 1 warm-up + 3 timed steps '
                                             f'({sec:.2f} s/step), torch-CPU fp32 restatement of the TF1 graph'}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def kernel_table(ph, n_rows, pk, xu=None):
    """Per-kernel-class roofline fractions of one training step (per-GPU rows), from the CUDA-event phase times.
    xu: measured MUFU rates (ops.probe_mufu) -> the sigmoid-throughput rows of the NADE kernels."""
    def row(name, ms, bound, work, peak, unit, note=''):
        if not ms:
            return None
        ach = work / (ms * 1e-3)
        return {'kernel': name, 'ms': ms, 'bound': bound, 'achieved': ach, 'peak': peak, 'unit': unit,
                'frac': ach / peak, 'note': note}
    rec_ms = ph.get('recur_fwd_ms', 0) + ph.get('recur_bwd_ms', 0)
    rows = [
        row('gemm_tc2 / gemm_tc (bf16 operand pairs)', ph.get('gemm_ms'), 'tensor', GEMM_TC_FLOPS_STEP * n_rows / 1e12, pk['tf_sust'],
            'TFLOP/s', 'algorithmic flops; x2-3 tf32 MMAs each'),
        row('lstm_tc2_fwd + lstm_tc3_bwd (pair recurrence)', rec_ms, 'tensor', 2 * RECUR_FLOPS_FWD * n_rows / 1e12,
            pk['tf_sust'], 'TFLOP/s', 'h.Wh and dG.Wh^T; latency chain per time step, see DESIGN.md'),
        row('nade_fwd (segment form)', ph.get('nade_fwd_ms'), 'fma', NADE_SEG_FLOPS_FWD * n_rows / 1e12, FP32_PEAK_TFLOPS,
            'TFLOP/s', 'SIMT: useful decode-dot flops vs the CUDA-core fp32 peak'),
        row('nade_bwd (segment form)', ph.get('nade_bwd_ms'), 'fma', 2 * NADE_SEG_FLOPS_FWD * n_rows / 1e12,
            FP32_PEAK_TFLOPS, 'TFLOP/s', 'SIMT: useful flops vs the CUDA-core fp32 peak'),
        row('nade_fwd HBM', ph.get('nade_fwd_ms'), 'hbm', n_rows * (1700 * 4 + M * D * 4 + 80 + 20) / 1e9, pk['hbm'], 'GB/s',
            'reads fc, writes d b_dec + nll'),
        # segment form: (1 + set bits among dims 0..D-2) * H sigmoids per (row, track): 5.15 * 256 * 5 at 5 % density; the
        # backward recomputes them. Peak = the measured rate of the kernels' own sigmoid (ex2 + rcp) on the XU pipe
        (row('nade_fwd XU', ph.get('nade_fwd_ms'), 'xu', n_rows * 5.15 * H * M / 1e9, xu['sigmoid'] / 2 / 1e9, 'Gsigmoid/s',
             'sigmoids of the segment form vs the measured MUFU rate')
         if xu and 'sigmoid' in xu else None),
        (row('nade_bwd XU', ph.get('nade_bwd_ms'), 'xu', n_rows * 5.15 * H * M / 1e9, xu['sigmoid'] / 2 / 1e9, 'Gsigmoid/s',
             'recomputed sigmoids vs the measured MUFU rate')
         if xu and 'sigmoid' in xu else None),
        row('pack (input staging)', ph.get('pack_ms'), 'hbm', n_rows * (420 + 1680 + 80) / 1e9, pk['hbm'], 'GB/s'),
        row('colsum (bias grads)', ph.get('colsum_ms'), 'hbm', n_rows * (2048 + 1024 + 1700) * 4 / 1e9, pk['hbm'], 'GB/s'),
    ]
    return [r for r in rows if r]


def profile_phases(model, x, args):
    """Device time of each phase of one training step, CUDA events on the current stream (max of 2 runs dropped,
    mean of the rest). 'dense_ms' = every GEMM-shaped phase (input projections, recurrences, Dense, weight grads)."""
    import torch
    from multinn_b200 import ops
    core = model._model
    gen = core.generators[0]
    marks = []

    def wrap(name, fn):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            marks.append((name, e0, e1))
            return r
        return f

    names = {'gemm': 'gemm', 'lstm_seq_fwd': 'recur_fwd', 'lstm_seq_bwd': 'recur_bwd', 'nade_logprob_fwd': 'nade_fwd',
             'nade_logprob_bwd': 'nade_bwd', 'pack_pianoroll': 'pack', 'clip_adam': 'optim', 'sqnorm_into': 'optim',
             'colsum': 'colsum', 'sum_into': 'loss_sum'}
    saved = {k: getattr(ops, k) for k in names}
    step = core.train_generators('adam', 0.01)
    try:
        for k, v in names.items():
            setattr(ops, k, wrap(v, saved[k]))
        step(x)
        torch.cuda.synchronize()
        marks.clear()
        step(x)
        torch.cuda.synchronize()
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
    acc = {}
    for name, e0, e1 in marks:
        acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
    n_gemm = sum(1 for name, _, _ in marks if name == 'gemm')
    acc = {k + '_ms': round(v, 3) for k, v in acc.items()}
    acc['gemm_launches'] = n_gemm
    acc['dense_ms'] = round(acc.get('gemm_ms', 0) + acc.get('recur_fwd_ms', 0) + acc.get('recur_bwd_ms', 0), 3)
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='C5', choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-batch', type=int, default=128, help='batch rows of the bounded CPU sample')
    ap.add_argument('--cpu-seconds', type=float, default=150.0, help='wall-clock bound of the reference arm\'s timed steps')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sampling', action='store_true')
    ap.add_argument('--batch', type=int, default=0, help='override the global batch (development aid)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
