# e2e step time vs number of tile batches of the three-stream pipeline (one GPU; TILES = shard size)
for nb in ${BATCHES:-8 12 16 24}; do
  python bench.py --no-flows --no-cpu-baseline --tiles ${TILES:-2048} --e2e-batches $nb 2>/dev/null > gpurun_out/e2e_$nb.json
  python - "$nb" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/e2e_{sys.argv[1]}.json"))
print("e2e_batches", sys.argv[1], "e2e ms", round(d["e2e"]["ms_per_step"], 2), "fp32 ms", round(d["e2e_fp32_heads"]["ms_per_step"], 2), "device ms", round(d["ms_per_step"], 2),
      "copies only", round(d["e2e"]["copies_only_ms_per_step"]["all_ranks_at_once"], 2))
PY
done
