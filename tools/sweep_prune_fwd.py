"""Forward time (3 kNN calls on real activations) for several prune window / soft-mark settings of the streaming kNN."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
grid = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(12, 180), (12, 164), (12, 172), (12, 188), (8, 180), (16, 180), (20, 188)]
for win, soft in grid:
    env = {**os.environ, "SEDNET_B200_SS_WIN": str(win), "SEDNET_B200_SS_SOFT": str(soft)}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_fwd.py"), "8", "3"], env=env, capture_output=True, text=True)
    print(win, soft, r.stdout.strip() or r.stderr[-300:], flush=True)
