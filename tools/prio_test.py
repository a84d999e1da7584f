import os, sys, time, subprocess
for prio in ("0", "1"):
    env = dict(os.environ, VSG_RANK_PRIORITY=prio)
    out = subprocess.run([sys.executable, "tools/sweep_one.py"], env=env, capture_output=True, text=True)
    print("priority", prio, out.stdout.strip(), out.stderr.strip()[-300:])
