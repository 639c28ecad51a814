#!/usr/bin/env python
"""Weight-ring diagnostics of the persistent frame kernel (library built with -DFK_FINE_MARKS): for every ring
stage of CTA 0, when the producer issued its TMA copy and when warp 0 started / stopped waiting for it."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LQT_B200_LIB", os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200_prof.so"))   # the build with timeline marks
from __graft_entry__ import load_package
load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms
spec = ms.spec_0p6b(0)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
eng = engine.Engine(mdir)
ids = engine.wrap_text_ids(ms.synthetic_text_ids(90, 1234))
prompt, trailing, pad = eng.build_prompt(ids, "en")
sp = eng.sampling(0.8, 50, 0.95, 3, 1234, 0)
eng.generate(prompt, trailing, pad, sp)
CAP = 200000
eng.timeline_arm(CAP, 0)
eng.generate(prompt, trailing, pad, sp)
import ctypes as C
out = np.zeros(CAP + 1, np.uint64)
# raw read of the whole buffer (consumer half + producer half)
n = eng.lib.lqt_debug_timeline(eng.h, 0, 0, out.ctypes.data_as(C.c_void_p), CAP)
pout = np.zeros(CAP + 1, np.uint64)
pn = eng.lib.lqt_debug_timeline(eng.h, 0, 1, pout.ctypes.data_as(C.c_void_p), CAP)
pclk = (pout[:pn] >> np.uint64(16)).astype(np.int64)
print("producer entries", pn)
print("consumer entries", n)
ent = out[:n]
clk = (ent >> np.uint64(16)).astype(np.int64); tag = (ent & np.uint64(0xffff)).astype(np.int64)
wb = {int(t & 0x3fff): int(c) for c, t in zip(clk, tag) if (t & 0xc000) == 0x8000}
we = {int(t & 0x3fff): int(c) for c, t in zip(clk, tag) if (t & 0xc000) == 0xc000}
waits = np.array([we[s] - wb[s] for s in sorted(wb) if s in we])
print(f"stages with a recorded wait: {len(waits)}; mean wait {waits.mean():.0f} cyc ({waits.mean()/1965:.2f} us), median {np.median(waits):.0f}, p90 {np.percentile(waits,90):.0f}, max {waits.max()}")
ks = sorted(wb)
lat = np.array([we[s] - int(pclk[s]) for s in ks if s in we and s < pn])
lead = np.array([wb[s] - int(pclk[s]) for s in ks if s in we and s < pn])
print(f"issue -> consumer wait begins (lead time): mean {lead.mean()/1965:.2f} us, median {np.median(lead)/1965:.2f} us, p10 {np.percentile(lead,10)/1965:.2f} us")
big = [s for s in ks if s in we and s < pn and we[s] - wb[s] > 1000]
print(f"stages where warp 0 waited > 1000 cycles: {len(big)} of {len(ks)}")
lat_big = np.array([we[s] - int(pclk[s]) for s in big])
if len(lat_big): print(f"  for those: issue -> landed: mean {lat_big.mean()/1965:.2f} us, median {np.median(lat_big)/1965:.2f} us, min {lat_big.min()/1965:.2f} us, max {lat_big.max()/1965:.2f} us")
i0 = 3500
for s in ks[i0:i0 + 48]:
    if s < pn: print(f"stage {s}: issued@{(int(pclk[s]) - int(pclk[ks[i0]]))/1965:8.2f} us  wait begin@{(wb[s] - int(pclk[ks[i0]]))/1965:8.2f}  landed/observed@{(we[s] - int(pclk[ks[i0]]))/1965:8.2f}  (waited {(we[s]-wb[s])/1965:.2f} us)")
eng.close()
