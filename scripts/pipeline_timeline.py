"""Debug aid: host + device timeline of TilePipeline.run (which stream does what when)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deepemia_b200 import engine, synthetic as syn

tiles_n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
protos = bench._prototypes()
tiles, offs, (proto_id, boxes, scores, classes) = bench.shard(0, 1, tiles_n)
d_probs = torch.as_tensor(protos, device=dev)[torch.as_tensor(proto_id, device=dev)].contiguous()
d_boxes, d_scores, d_classes = (torch.as_tensor(a, device=dev) for a in (boxes, scores, classes))
pw = engine.pitch_words_for(1024)
arena = torch.empty((min(len(boxes), 250000), 1024, pw), dtype=torch.int32, device=dev)
pipe = engine.TilePipeline(1024, 1024, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7, frames=arena, batches=batches,
                           paste_ctas_per_sm=ctas, device=dev)
marks = []
orig = {}
def wrap(name):
    f = getattr(engine, name); orig[name] = f
    def g(*a, **k):
        st = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True); e0.record(st)
        h0 = time.perf_counter()
        r = f(*a, **k)
        h1 = time.perf_counter()
        e1 = torch.cuda.Event(enable_timing=True); e1.record(st)
        marks.append((name, h0, h1, e0, e1))
        return r
    setattr(engine, name, g)
for nm in ("paste", "trace", "dedup_smart", "overlap_rules", "containment_rules", "measure_list", "paste_plan"):
    wrap(nm)
for it in range(4):
    marks.clear()
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True); base.record()
    h_base = time.perf_counter()
    res = pipe.run(d_probs, d_boxes, d_scores, d_classes, offs)
    h_end = time.perf_counter()
    end = torch.cuda.Event(enable_timing=True); end.record()
    torch.cuda.synchronize()
    if it == 3:
        print(f"step: device {base.elapsed_time(end):.2f} ms, host returned after {(h_end - h_base) * 1e3:.2f} ms")
        for name, h0, h1, e0, e1 in marks:
            print(f"{name:18s} host {1e3 * (h0 - h_base):7.2f} -> {1e3 * (h1 - h_base):7.2f}   dev {base.elapsed_time(e0):7.2f} -> {base.elapsed_time(e1):7.2f}")
