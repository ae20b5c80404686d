"""Debug: Composer training step time at a small per-GPU batch under different time-chunk pipeline settings.
python tools/pipeline_bench.py [B] -- every configuration runs in a fresh subprocess (the settings are class attributes read
from the environment at import)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = [
    ('wavefront only (no pipeline)', {'MNN_PIPE_MAX_BATCH': '0'}),
    ('pipeline default', {}),
    ('fwd 32,16', {'MNN_PIPE_FWD_BUDGETS': '32,16'}),
    ('fwd 32,16 slow 3', {'MNN_PIPE_FWD_BUDGETS': '32,16', 'MNN_PIPE_SLOW_HOOKS': '3'}),
    ('fwd 32,16 slow 4', {'MNN_PIPE_FWD_BUDGETS': '32,16', 'MNN_PIPE_SLOW_HOOKS': '4'}),
    ('bwd 48,24', {'MNN_PIPE_BWD_BUDGETS': '48,24'}),
    ('fwd 32,16 slow 3 + bwd 48,24', {'MNN_PIPE_FWD_BUDGETS': '32,16', 'MNN_PIPE_SLOW_HOOKS': '3', 'MNN_PIPE_BWD_BUDGETS': '48,24'}),
    ('default slow 3', {'MNN_PIPE_SLOW_HOOKS': '3'}),
    ('default slow 1', {'MNN_PIPE_SLOW_HOOKS': '1'}),
    ('full wgrads 3', {'MNN_PIPE_FULL_WGRADS': '3'}),
    ('full wgrads 1', {'MNN_PIPE_FULL_WGRADS': '1'}),
]

if __name__ == '__main__':
    B = sys.argv[1] if len(sys.argv) > 1 else '256'
    only = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else None
    for i, (name, env) in enumerate(CONFIGS):
        if only is not None and i not in only:
            continue
        e = dict(os.environ)
        e.update(env)
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--batch', B, '--steps', '5', '--warmup', '3',
                                  '--no-cpu', '--no-sampling'], env=e, capture_output=True, text=True, timeout=300)
            line = [l for l in out.stdout.splitlines() if l.startswith('{')]
            if not line:
                print(f'{i} {name}: FAILED rc={out.returncode} {out.stderr[-400:]}', flush=True)
                continue
            j = json.loads(line[-1])
            print(f'{i} {name}: {j["ms_per_step"]:.3f} ms/step, {j["value"] / 1e6:.3f} M time-steps/s, e2e {j["e2e"]["ms_per_step"]:.3f} ms, '
                  f'loss {j.get("final_loss")}', flush=True)
        except subprocess.TimeoutExpired:
            print(f'{i} {name}: TIMEOUT', flush=True)
