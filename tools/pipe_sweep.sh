run() { # label, env...
  label=$1; shift
  env "$@" timeout 120 python bench.py --batch ${BATCH:-256} --no-cpu --no-sampling --steps 10 --warmup 3 > /tmp/sw.log 2>&1
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("/tmp/sw.log").read().strip().splitlines()[-1])
    ph=d["roofline"]["phases_ms"]
    print(sys.argv[1], round(d["ms_per_step"],3), "fwd",ph.get("recur_fwd_ms"),"bwd",ph.get("recur_bwd_ms"),"gemm",ph.get("gemm_ms"), flush=True)
except Exception as e:
    print(sys.argv[1], "FAILED", open("/tmp/sw.log").read()[-300:])
PY
}
run base X=1
run b64_64 MNN_PIPE_FWD_BUDGETS=64,64
run b128_64 MNN_PIPE_FWD_BUDGETS=128,64
run b64_32 MNN_PIPE_FWD_BUDGETS=64,32
run hooks1 MNN_PIPE_SLOW_HOOKS=1
run hooks3 MNN_PIPE_SLOW_HOOKS=3
run hooks4 MNN_PIPE_SLOW_HOOKS=4
run chunk64 MNN_WAVEFRONT_CHUNK=64
run chunk16 MNN_WAVEFRONT_CHUNK=16
run b64_64_h3 MNN_PIPE_FWD_BUDGETS=64,64 MNN_PIPE_SLOW_HOOKS=3
run nopipe MNN_PIPE_MAX_BATCH=0 MNN_PIPE_MAX_BATCH_FWD=0
run nowave MNN_PIPE_MAX_BATCH=0 MNN_PIPE_MAX_BATCH_FWD=0 MNN_WAVEFRONT_MAX_BATCH=0
