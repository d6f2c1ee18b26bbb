"""Two steps of the BASELINE configs[1] workload (64 clips x 10 s, bf16) for ncu: the first warms up, the second is the one to read.
Usage (on the GPU box):  ncu ... python profiles/prof_step.py [clips] [seconds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200.smoke import build_models  # noqa: E402
from bench import make_inputs, n_frames  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
n = int(seconds * 16000)
T = n_frames(n)
m = build_models("bf16")
inp = {k: v.cuda() for k, v in make_inputs(clips, n, T, seed=1000).items()}
for it in range(2):
    if it == 1:   # with `ncu --profile-from-start off` only the second step is seen (launch indices then count from its first kernel)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    v = m.predict_from_embeddings(inp["audio"], inp["emo"])
    fv = m.convert_coeff2verts(inp["coeff"], inp["pose"], inp["shape"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(v.shape), tuple(fv.shape))
