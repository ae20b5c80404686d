"""Runs every BASELINE.json configuration at its FULL shape for a few training steps (and a short generation) on one
GPU and prints one JSON line per configuration: ms per step, time-steps/s, loss trajectory (finite; falling for the
NLL-trained modes), peak memory. The five configs are parity-test cases, not bench lines (bench.py measures the headline
Composer workload); this tool answers "does each of them run at size, and how fast" in a single gpurun call.

Hyper-parameters not fixed by BASELINE.json are the reference defaults (H = 256, R = [512,256], keep 0.9;
Feedback-RNN R = [256,256] with a [256,128] feedback LSTM; DBN [168,84]; Joint RBM k = 10).
C4 is a data-parallel 8-GPU config: on one GPU its per-GPU shard [128,256,84,5] is run (pass --full-c4 for all 1024 rows).

Usage (GPU box):  python tools/config_check.py [C1 C2 C3 C4 C5] [--steps 3] [--full-c4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {
    'C1': dict(mode='jamming', shape=(64, 64), kw=dict(encoder='Pass', generator='NADE')),
    'C2': dict(mode='composer', shape=(256, 128), kw=dict(encoder='Pass', generator='NADE')),
    'C3': dict(mode='joint', shape=(512, 128), kw=dict(encoder='DBN', encoder_hidden=[168, 84], generator='RBM')),
    'C4': dict(mode='feedback-rnn', shape=(1024, 256),
               kw=dict(encoder='DBN', encoder_hidden=[168, 84], generator='NADE', num_hidden_rnn=(256, 256),
                       feedback=[256, 128])),
    'C5': dict(mode='composer', shape=(2048, 256), kw=dict(encoder='Pass', generator='NADE')),
}


def run(name, steps, full_c4):
    import numpy as np
    import torch
    from multinn_b200.multinn import MultINN, default_config, default_params
    cfg = CONFIGS[name]
    B, T = cfg['shape']
    note = ''
    if name == 'C4' and not full_c4:
        B, note = B // 8, 'per-GPU shard of the 8-GPU data-parallel config'
    torch.cuda.reset_peak_memory_stats()
    model = MultINN(default_config(), default_params(mode=cfg['mode'], keep_prob=0.9, **cfg['kw']), cfg['mode'])
    rng = np.random.default_rng(23)
    x = torch.from_numpy((rng.random((B, T, 84, 5)) < 0.05).astype(np.uint8)).cuda()
    step = model.train_generators('adam', 0.01)
    losses = [float(step(x))]                                    # warm-up step (allocations, tensor maps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    dev_losses = [step(x).clone() for _ in range(steps)]       # the step returns a view of a reused device buffer
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    losses += [float(l) for l in dev_losses]
    out = {'config': name, 'mode': cfg['mode'], 'shape': [B, T, 84, 5], 'note': note, 'ms_per_step': round(ms, 3),
           'time_steps_per_s': round(B * T / (ms * 1e-3)), 'losses': [round(l, 5) for l in losses],
           'finite': bool(np.all(np.isfinite(losses)))}
    if cfg['kw']['generator'] == 'NADE':
        out['falling'] = bool(losses[-1] < losses[0])
    intro = x[:min(B, 64), :32].contiguous()
    s = model.generate(intro, 8)
    out['generated'] = list(s.shape)
    out['generated_binary'] = bool(((s == 0) | (s == 1)).all())
    out['peak_mem_gb'] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    print(json.dumps(out), flush=True)
    del model, step, x
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='*', default=list(CONFIGS))
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--full-c4', action='store_true')
    args = ap.parse_args()
    ok = True
    for name in args.configs:
        try:
            r = run(name, args.steps, args.full_c4)
            ok &= r['finite'] and r['generated_binary']
        except Exception as e:      # noqa: BLE001 - report and go on to the next configuration
            ok = False
            import traceback
            print(json.dumps({'config': name, 'error': repr(e)[:400], 'traceback': traceback.format_exc()[-1500:]}),
                  flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
