"""Development: probe_bin.py under several environment settings (one process each); prints the MEAN line of every run.
usage: sweep_env.py n_rec k "VAR=a,VAR2=b" "VAR=c" ..."""
import os, subprocess, sys
n, k = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
for cfg in sys.argv[3:] or [""]:
    env = dict(os.environ)
    for kv in cfg.split(","):
        if "=" in kv:
            a, b = kv.split("=", 1)
            env[a] = b
    r = subprocess.run([sys.executable, os.path.join(here, "probe_bin.py"), n, k, "4"], env=env, capture_output=True, text=True)
    mean = [l for l in r.stdout.splitlines() if l.startswith("MEAN")]
    print("%-60s %s" % (cfg or "(default)", mean[0][5:] if mean else "FAILED " + r.stderr[-300:]), flush=True)
