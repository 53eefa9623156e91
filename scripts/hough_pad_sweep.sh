# experiment: frames in flight of the Hough kernel (shared-memory padding per CTA) vs step time
for pad in ${PADS:-0 26 46 84}; do
  EMIA_HOUGH_SMEM_PAD_KB=$pad python bench.py --flows-only scalebar --no-cpu-baseline 2>/dev/null > gpurun_out/hp_$pad.json
  python - "$pad" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/hp_{sys.argv[1]}.json"))["scalebar"]
print("pad_kb", sys.argv[1], "frames/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "parity", d["parity_vs_gpu_on_sample"])
PY
done
