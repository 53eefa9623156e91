"""torchrun script (N >= 2 GPUs): BASELINE config 3 in miniature — one micrograph split over the ranks, one NCCL all-gather of the
bit-packed instances, identical global de-dup on every rank — must reproduce the golden result of the reference's
tile_based_inference_pipeline (tests/golden/flows_golden.npz, case "tile_pipeline").
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_split_micrograph.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))
import flow_cases  # noqa: E402
from deepemia_b200 import distributed as D, synthetic as syn  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gold = np.load(os.path.join(HERE, "golden", "flows_golden.npz"))
    for name in ("tile_pipeline", "tile_pipeline_no_edge_filter"):
        case = flow_cases.CASES[name]
        image = flow_cases.make_image(case["image_seed"], *case["shape"])
        pred = syn.FakeHeadPredictor(**case["predictors"][0])
        masks, scores, classes = D.split_micrograph_pipeline(pred, image, *case["args"], **case["kwargs"])
        n = len(gold[f"{name}/scores"])
        h, w = (int(v) for v in gold[f"{name}/shape"])
        ref = np.unpackbits(gold[f"{name}/bits"], axis=1)[:, :h * w].reshape(n, h, w).astype(bool)
        assert len(masks) == n, (len(masks), n)
        for i in range(n):
            assert np.array_equal(np.asarray(masks[i]) != 0, ref[i]), f"rank {dist.get_rank()} {name} mask {i}"
            assert float(scores[i]) == float(gold[f"{name}/scores"][i]) and int(classes[i]) == int(gold[f"{name}/classes"][i])
            assert np.asarray(masks[i]).dtype.kind == str(gold[f"{name}/kinds"][i])
        print(f"rank {dist.get_rank()}/{dist.get_world_size()}: {name} OK ({n} instances, predictor calls on this rank: {pred.calls})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
