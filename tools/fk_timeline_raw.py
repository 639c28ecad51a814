import sys, os
sys.path.insert(0, '/root/repo')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("LQT_B200_LIB", os.path.join(ROOT, "leaxer-qwen3-tts_b200", "csrc", "liblqt_b200_prof.so"))   # the build with timeline marks
from __graft_entry__ import load_package
load_package()
from leaxer_qwen3_tts_b200 import engine, modelspec as ms
spec = ms.spec_0p6b(0)
mdir = ms.generate_model_dir(ms.default_model_dir(spec), spec)
eng = engine.Engine(mdir)
ids = engine.wrap_text_ids(ms.synthetic_text_ids(90, 1234))
prompt, trailing, pad = eng.build_prompt(ids, "en")
sp = eng.sampling(0.8, 50, 0.95, 6, 1234, 0)
eng.generate(prompt, trailing, pad, sp)
eng.timeline_arm(400000, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
eng.generate(prompt, trailing, pad, sp)
clk, tag = eng.timeline_read()
KINDS = {1: "A", 2: "B", 3: "C", 4: "D", 5: "E", 6: "head", 7: "sample", 8: "inproj"}
n = len(clk)
start = n - (int(sys.argv[2]) if len(sys.argv) > 2 else 400)
for i in range(start, start + 150):
    st, kind, pt = (tag[i] >> 9) & 1, (tag[i] >> 4) & 31, tag[i] & 15
    print(f"{'cp' if st else 'tk'} {KINDS.get(kind, kind):6s} pt{pt:2d}  +{int(clk[i]) - int(clk[i-1]):6d} cyc")
eng.close()
