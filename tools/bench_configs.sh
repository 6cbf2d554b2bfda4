set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['config']['workload'][:60], '| value %.0f (%.4f ms/step) e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']),
          {k: (round(v['value']), round(v['ms_per_step'], 4)) for k, v in d['modes'].items() if isinstance(v, dict)},
          'parity', d['parity']['mismatches'], 'status', d['status_bits'], {m: {k: round(v['ms'], 3) for k, v in t.items()} for m, t in d['kernels'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 400 python bench.py --workload C3 --seq 64 --no-cpu-baseline --e2e-reps 2 > gpurun_out/bench_c3_s64.json 2> gpurun_out/bench_c3_s64.err; tail -2 gpurun_out/bench_c3_s64.err; summ gpurun_out/bench_c3_s64.json
timeout 500 python bench.py --workload C3 --seq 256 --steps 8 --warmup 3 --no-cpu-baseline --e2e-reps 1 --reps 3 --parity-seqs 2 > gpurun_out/bench_c3_s256.json 2> gpurun_out/bench_c3_s256.err; tail -2 gpurun_out/bench_c3_s256.err; summ gpurun_out/bench_c3_s256.json
timeout 600 python bench.py --workload C4 --seq 64 --no-cpu-baseline --e2e-reps 2 > gpurun_out/bench_c4_s64.json 2> gpurun_out/bench_c4_s64.err; tail -2 gpurun_out/bench_c4_s64.err; summ gpurun_out/bench_c4_s64.json
timeout 400 python bench.py --workload C2 --seq 128 --steps 20 --warmup 4 --no-cpu-baseline --e2e-reps 2 > gpurun_out/bench_c5synth_s128.json 2> gpurun_out/bench_c5synth_s128.err; tail -2 gpurun_out/bench_c5synth_s128.err; summ gpurun_out/bench_c5synth_s128.json
timeout 200 python tools/bench_k2.py > gpurun_out/k2_micro.txt 2>&1; cat gpurun_out/k2_micro.txt
