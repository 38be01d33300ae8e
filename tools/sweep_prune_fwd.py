"""Forward time (3 kNN calls on real activations) for several prune window / soft-mark settings of the streaming kNN."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for win, soft in ((24, 124), (12, 124), (12, 140), (8, 140), (16, 132), (12, 148), (8, 148), (4, 148)):
    env = {**os.environ, "SEDNET_B200_SS_WIN": str(win), "SEDNET_B200_SS_SOFT": str(soft)}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_fwd.py"), "8", "3"], env=env, capture_output=True, text=True)
    print(win, soft, r.stdout.strip() or r.stderr[-300:], flush=True)
