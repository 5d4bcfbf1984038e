"""Reference arm under different OpenMP settings (run on the GPU box: its host cores are the baseline's)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for env in ({}, {"OMP_WAIT_POLICY": "active"}, {"OMP_WAIT_POLICY": "active", "OMP_PROC_BIND": "true"}, {"OMP_PROC_BIND": "spread", "OMP_PLACES": "cores"}):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "10", "--warmup", "2"],
                       capture_output=True, text=True, env=dict(os.environ, **env))
    j = json.loads(r.stdout.strip().splitlines()[-1])
    print(env, f"{j['value']:.4g} evals/s", f"{j['ms_per_step']:.1f} ms/step", flush=True)
