import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch
sys.path.insert(0, %r)
from vidseg_diffusion_b200.linear import gemm_split, split
a = torch.randn(256, 128, device="cuda"); w = torch.randn(128, 128, device="cuda") / 11.3
out, _ = gemm_split(split(a), split(w))
torch.cuda.synchronize()
want = a.double() @ w.double().T
print("rel err", float((out.double() - want).abs().max() / want.abs().max()))
''' % ROOT
for fmt in ("1", "2", "0"):
    env = dict(os.environ, VIDSEG_DEBUG_FMT=fmt, CUDA_LAUNCH_BLOCKING="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("fmt", fmt, "rc", r.returncode, r.stdout.strip()[-200:], r.stderr.strip()[-300:].replace("\n", " | "))
